"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/phylomap_b200.h declares, the
host-only entries work, and the compute entries refuse to run without a B200 (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import cases
import phylomap_b200 as pb
from phylomap_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "phylomap_b200.h")).read()
    declared = set(re.findall(r"\b(pm_[A-Za-z0-9_]+)\s*\(", hdr)) - {"pm_allreduce_fn"}
    assert declared, "no declarations found"
    L = capi.lib()
    for name in sorted(declared):
        assert hasattr(L, name), "%s declared in the header but not exported" % name
    assert declared == set(capi.EXPORTS)


def test_ncols_layouts():
    L = capi.lib()
    assert L.pm_ncols(capi.PM_V_PLAIN, 2) == 4          # man/sumstatMCMC.Rd:18
    assert L.pm_ncols(capi.PM_V_BIGTREE, 4) == 16
    assert L.pm_ncols(capi.PM_V_BF, 2) == 9             # R/sumstatMCMCbf.R:33
    assert L.pm_ncols(capi.PM_V_MT, 2) == 9
    assert L.pm_ncols(capi.PM_V_KS, 4) == 4 + 16 + 2 + 3 + 1
    assert L.pm_ncols(capi.PM_V_KSMT, 6) == 6 + 36 + 2 + 6 + 1
    assert L.pm_ncols(capi.PM_V_DIC2S, 2) == 10        # src/phylomap.cpp:3233
    assert L.pm_ncols(capi.PM_V_DICKS, 4) == 4 + 16 + 2 + 3 + 2


@pytest.mark.parametrize("T,seed", [(2, 1), (3, 2), (17, 3), (200, 4), (3000, 5)])
def test_tree_order_is_a_valid_pruningwise_order(T, seed):
    t = synth.yule_tree(T, seed)
    nen, nodelist, root = t.order()
    E = t.E
    assert sorted(nen) == list(range(1, E + 1))
    parent, child = t.edge[:, 0], t.edge[:, 1]
    assert root == T + 1 and root not in child
    done = set(range(1, T + 1))
    for i in range(T - 1):
        a, b = nen[2 * i] - 1, nen[2 * i + 1] - 1
        assert parent[a] == parent[b]                      # sibling pairs adjacent (src/phylomap.cpp:508-510)
        assert child[a] in done and child[b] in done       # children before parents
        done.add(parent[a])
    assert parent[nen[-1] - 1] == root                     # myreorder: last edge's parent is the root
    assert len(nodelist) == T - 2 and len(set(nodelist)) == T - 2
    seen = {root}
    pe = {child[e]: parent[e] for e in range(E)}
    for v in nodelist:                                     # makenodelist: top-down
        assert pe[v] in seen
        seen.add(v)
    # makenodelist reads the parents of the pruning-wise edge pairs backwards (R/sumstatMCMC.R:11-17)
    assert [parent[nen[E - 2 * i - 1] - 1] for i in range(1, T - 1)] == list(nodelist)


def test_tree_order_rejects_non_binary():
    edge = np.asfortranarray(np.array([[4, 1], [4, 2], [4, 3]], dtype=np.int32))
    err = C.create_string_buffer(256)
    nen = np.zeros(3, dtype=np.int32)
    nl = np.zeros(1, dtype=np.int32)
    root = C.c_int32()
    rc = capi.lib().pm_tree_order(capi.ptr(edge), 3, 3, capi.ptr(nen), capi.ptr(nl), C.byref(root), err, 256)
    assert rc == capi.PM_ERR_ARG and b"binary" in err.value


def test_no_cpu_fallback():
    """Without a usable sm_100 device every sampler entry fails with PM_ERR_CUDA."""
    if capi.lib().pm_device_count() > 0:
        pytest.skip("a B200 is present")
    z = cases.tree2(T=8)
    with pytest.raises(capi.PhylomapError) as e:
        pb.sumstatMCMC(z, cases.Q2, cases.PID2, 0.2, 2)
    assert e.value.code == capi.PM_ERR_CUDA


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "phylomap_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in src.replace("oracle/bridge.py takes (tests only)", "") or f == "tree.py"
                assert "import oracle" not in src and "from oracle" not in src


def _probe(seed, kind, n, a=0.0, b=0.0):
    out = np.zeros(n)
    capi.lib().pm_rng_probe(seed, kind, n, a, b, capi.ptr(out))
    return out


def test_host_rng_matches_known_r_outputs_and_the_oracle(oracle):
    """The product's own restatement of R's RNG (pm_rrng.hpp, used by the rate updates) against published R outputs
    and, draw for draw, against the oracle's independent restatement."""
    np.testing.assert_allclose(_probe(1, 0, 3), [0.2655087, 0.3721239, 0.5728534], atol=5e-8)      # set.seed(1); runif(3)
    np.testing.assert_allclose(_probe(123, 0, 3), [0.2875775, 0.7883051, 0.4089769], atol=5e-8)
    np.testing.assert_allclose(_probe(1, 1, 3), [0.7551818, 1.1816428, 0.1457067], atol=5e-8)      # rexp(3)
    np.testing.assert_allclose(_probe(1, 2, 3), [-0.6264538, 0.1836433, -0.8356286], atol=5e-8)    # rnorm(3)
    for kind, name in [(0, "unif"), (1, "exp"), (2, "norm")]:
        assert np.array_equal(_probe(77, kind, 2000), oracle.rng_probe(77, name, 2000))
    for shape, scale in [(0.3, 2.0), (1.0, 1.0), (2.5, 0.5), (7.0, 0.1), (20.0, 0.05)]:
        assert np.array_equal(_probe(9, 3, 3000, shape, scale), oracle.rng_probe(9, "gamma", 3000, shape, scale))


def _expect_arg_error(fn, fragment):
    with pytest.raises(capi.PhylomapError) as e:
        fn()
    assert e.value.code == capi.PM_ERR_ARG, e.value
    assert fragment in e.value.msg, e.value.msg


def test_malformed_inputs_are_reported_before_any_device_work():
    """Host-side validation of the reference's tree inputs (x$edge, nen, nodelist, root, maps): PM_ERR_ARG with a
    message, also on a machine without a GPU."""
    z = cases.tree2(T=8)
    nen, nodelist, root = z.order()
    B = np.asfortranarray(np.eye(2) + cases.Q2 / 0.2)
    call = lambda **kw: pb.maketreelistMCMC(kw.get("z", z), cases.Q2, cases.PID2, B, kw.get("Om", 0.2), kw.get("nen", nen),
                                            kw.get("nodelist", nodelist), kw.get("root", root), 3)
    _expect_arg_error(lambda: call(nen=nen[::-1].copy()), "pruning-wise")            # parents before children
    bad = nen.copy(); bad[1] = bad[0]
    _expect_arg_error(lambda: call(nen=bad), "nen")
    swapped = nen.copy(); swapped[[1, 2]] = swapped[[2, 1]]
    _expect_arg_error(lambda: call(nen=swapped), "sibling")
    _expect_arg_error(lambda: call(nodelist=nodelist[::-1].copy()), "top-down")
    _expect_arg_error(lambda: call(root=1), "root")
    _expect_arg_error(lambda: call(Om=0.0), "Omega")
    neg = pb.PhyloTree(z.edge, z.edge_length, z.states, [m.copy() for m in z.maps], z.mapnames)
    neg.maps[0][0] = -1.0
    _expect_arg_error(lambda: call(z=neg), "segment length")
    edge = z.edge.copy(); edge[0, 0] = 1                                                # a tip as a parent
    _expect_arg_error(lambda: call(z=pb.PhyloTree(edge, z.edge_length, z.states, z.maps, z.mapnames)), "edge")
    with pytest.raises(capi.PhylomapError) as e:                                         # bf is 2-state only
        pb.sumstatMCMCbf(cases.tree_n(cases.jc(3), T=6), cases.jc(3), np.full(3, 1 / 3), 1.0, 2, cases.PRIOR_BF)
    assert e.value.code == capi.PM_ERR_ARG
    with pytest.raises(capi.PhylomapError) as e:                                         # prior too short
        pb.sumstatMCMCks(cases.tree_hidden(cases.q4(), T=6), cases.q4(), np.full(4, .25), 4.0, 2, [1.0, 2.0])
    assert e.value.code == capi.PM_ERR_ARG


def test_sumstatEXP_rejects_complex_spectrum():
    """R/sumstatEXP.R:26-29 passes eigen(Q) as real matrices; a generator with complex eigenvalues cannot be expressed."""
    import phylomap_b200 as pb
    from phylomap_b200 import synth
    Q = np.array([[-1.0, 1.0, 0.0], [0.0, -1.0, 1.0], [1.0, 0.0, -1.0]])   # cyclic: eigenvalues -1.5 +- 0.87i
    t = synth.yule_tree(4, 1).with_states(np.array([1, 2, 3, 1], dtype=np.int32))
    with pytest.raises(ValueError):
        pb.sumstatEXP(t, Q, np.full(3, 1 / 3), 2)


def test_column_names_follow_the_reference():
    import phylomap_b200 as pb
    assert pb.colnames(pb.sumstatMCMCbf) == ["time 0", "time 1", "n00", "n01", "n10", "n11", "l01", "l10", "root_state"]
    assert pb.colnames("maketreelistMCMCmt")[-1] == "tree_number"
    two, four = pb.colnames(pb.sumstatMCMC2sDICt), pb.colnames(pb.sumstatMCMCksDICt, 4)
    assert len(two) == capi.lib().pm_ncols(capi.PM_V_DIC2S, 2) and two[6:8] == ["l01", "l10"] and two[-1] == "log(p(y|Q))"
    assert len(four) == capi.lib().pm_ncols(capi.PM_V_DICKS, 4) and four[20:25] == ["l01", "l10", "k01", "k10", "gamma"]
    assert len(pb.colnames(pb.sumstatMCMCksDICt, 6)) == capi.lib().pm_ncols(capi.PM_V_DICKS, 6)
    with pytest.raises(KeyError):
        pb.colnames(pb.sumstatMCMC)


def _ladder(T):
    edge, node = [], T + 1
    for t in range(T, 2, -1):
        edge.append((node, t)); edge.append((node, node + 1)); node += 1
    edge.append((node, 1)); edge.append((node, 2))
    return pb.PhyloTree(np.array(edge, dtype=np.int32), np.ones(len(edge)))


@pytest.mark.parametrize("shape,T", [("yule", 2), ("yule", 3), ("yule", 9), ("yule", 500), ("yule", 10000), ("ladder", 40),
                                     ("ladder", 700), ("balanced", 8), ("balanced", 1024)])
def test_clade_schedules_are_valid(shape, T):
    """Host logic behind k_prune_clade / k_nodes_clade (pm_tree.hpp::build_clade_schedule), checked by the library's own
    verifier for several tree shapes and clade sizes: every internal node once, dependencies respected, flags consistent."""
    from phylomap_b200 import synth
    t = _ladder(T) if shape == "ladder" else synth.balanced_tree(int(np.log2(T))) if shape == "balanced" else synth.yule_tree(T, seed=T)
    nen, nodelist, root = t.order()
    edge = np.asfortranarray(t.edge)
    for clade_max in (1, 8, max(8, (T - 1) // 64), 10 ** 6):
        stats = np.zeros(8, dtype=np.int64)
        err = C.create_string_buffer(512)
        rc = capi.lib().pm_debug_clade_schedule(capi.ptr(edge), t.E, t.T, capi.ptr(nen), capi.ptr(nodelist), int(root), 8, clade_max,
                                                capi.ptr(stats), err, 512)
        assert rc == 0, err.value.decode()
        n1, ntop, nlev, lo, hi, nprev, d1, dtop = stats.tolist()
        assert n1 + ntop == T - 1 and d1 + dtop == T - 2
        if clade_max >= T:                      # one clade: everything in one warp's sequence, nothing on top
            assert ntop == 0 and hi == T - 1
        if shape == "yule" and T == 10000 and clade_max == (T - 1) // 64:
            assert ntop < 400 and hi - lo <= 0.02 * hi and nprev > 0.6 * (T - 1) * 0.66   # the benchmark's schedule: balanced, shallow top


def test_read_newick_numbers_nodes_like_ape():
    """ape::read.tree numbering: tips 1..T in order of appearance, the root T + 1, internal nodes in preorder; edges in
    preorder (checked against the package's 3 951-tip tree in tests/test_rds.py where the reference is present)."""
    from phylomap_b200 import synth
    t = synth.read_newick("((A:1,B:2):0.5,(C:3,('D d':4,E:5):1.5):2.5);")
    assert t.tip_label == ["A", "B", "C", "D d", "E"]
    assert t.edge.tolist() == [[6, 7], [7, 1], [7, 2], [6, 8], [8, 3], [8, 9], [9, 4], [9, 5]]
    np.testing.assert_allclose(t.edge_length, [0.5, 1, 2, 2.5, 3, 1.5, 4, 5])
    nen, nodelist, root = t.order()
    assert root == 6 and sorted(nen.tolist()) == list(range(1, 9))
    with pytest.raises(ValueError):
        synth.read_newick("(A:1,B:1,C:1);")          # not binary

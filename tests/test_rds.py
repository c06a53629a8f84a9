"""RDS reader / writer (phylomap_b200/rds.py): the container of the reference's tree fixture and of the traces its DIC
helper saves (R/Squamate_tree_setup.R:85, R/sourceme.R:533)."""
import os

import numpy as np
import pytest

import cases
import phylomap_b200 as pb
from phylomap_b200 import rds

REF_TREE = "/root/reference/inst/extdata/Squamate/phylomap_compatible_squamate_tree.RData"
REF_DIC = "/root/reference/inst/extdata/Squamate/Squamate_DIC_AIC_results.RData"


def test_round_trip(tmp_path):
    obj = {"edge": np.arange(12, dtype=np.int32).reshape(6, 2), "edge.length": np.linspace(0.1, 0.6, 6),
           "maps": [np.array([0.05, 0.05]), np.array([0.2])], "tip.label": ["a", "b", None, "d"],
           "flag": np.array([True, False]), "nested": {"x": np.array([1.5]), "y": ["s"]}}
    for compress in (True, False):
        p = str(tmp_path / ("t%d.rds" % compress))
        rds.write_rds(p, obj, compress=compress)
        back = rds.read_rds(p)
        assert list(back) == list(obj)                       # names, in order
        np.testing.assert_array_equal(back["edge"], obj["edge"])   # dim restored column-major
        assert back["edge"].dtype == np.int32
        np.testing.assert_array_equal(back["edge.length"], obj["edge.length"])
        np.testing.assert_array_equal(back["maps"][0], obj["maps"][0])
        assert back["tip.label"] == obj["tip.label"] and back["nested"]["y"] == ["s"]
        assert back["flag"].dtype == bool and back["flag"].tolist() == [True, False]


def test_trace_with_column_names(tmp_path):
    """A sampler trace saved the way make_12_chains does (colnames + saveRDS) reads back with its dimnames."""
    mat = np.arange(30, dtype=np.float64).reshape(3, 10)
    names = ["t0", "t1", "n00", "n01", "n10", "n11", "l01", "l10", "root_state", "log(p(y|Q))"]
    p = str(tmp_path / "trace.rds")
    rds.write_rds(p, mat, colnames=names)
    back = rds.read_rds(p)
    assert isinstance(back, rds.RObject) and back.attributes["dimnames"][1] == names
    np.testing.assert_array_equal(back.value, mat)


def test_tree_round_trip(tmp_path):
    z = cases.tree2(T=9, S=1, seed=2)
    p = str(tmp_path / "tree.rds")
    rds.write_rds(p, z.to_mapping())
    raw = rds.read_rds(p)
    assert raw.attributes["class"] == ["phylo"]
    back = pb.PhyloTree.read_rds(p)
    np.testing.assert_array_equal(back.edge, z.edge)
    np.testing.assert_array_equal(back.edge_length, z.edge_length)
    np.testing.assert_array_equal(back.states, z.states)
    for a, b in zip(back.maps, z.maps):
        np.testing.assert_array_equal(a, b)
    for a, b in zip(back.mapnames, z.mapnames):
        np.testing.assert_array_equal(a, b)


def test_squamate_fixture_shape():
    z = cases.squamate_tree()
    assert (z.T, z.E) == (3951, 7900) and all(len(m) == 100 for m in z.maps)
    np.testing.assert_allclose(z.edge_length.sum(), 87740.484341, rtol=1e-9)
    np.testing.assert_allclose([m.sum() for m in z.maps], z.edge_length, rtol=1e-12)
    nen, nodelist, root = z.order()
    assert root == 3952 and len(nen) == 7900 and len(nodelist) == 3949


@pytest.mark.skipif(not os.path.exists(REF_TREE), reason="the reference tree is only present in the build container")
def test_reads_the_reference_files():
    """readRDS of the package's own fixtures; the derived npz fixture is what the file holds."""
    raw = rds.read_rds(REF_TREE)
    assert raw.attributes == {"class": ["phylo"], "order": ["cladewise"]}
    z, f = pb.PhyloTree.from_mapping(raw), cases.squamate_tree()
    np.testing.assert_array_equal(z.edge, f.edge)
    np.testing.assert_array_equal(z.edge_length, f.edge_length)
    np.testing.assert_array_equal(z.states, f.states)
    np.testing.assert_array_equal(np.concatenate(z.maps), np.concatenate(f.maps))
    np.testing.assert_array_equal(np.concatenate(z.mapnames), np.concatenate(f.mapnames))
    assert len(z.tip_label) == 3951 and z.tip_label[0] == "Sphenodon_punctatus"
    dic = rds.plain(rds.read_rds(REF_DIC))
    assert sorted(dic) == ["AIC", "diffuse", "restricted", "spike"] and dic["restricted"].tolist() == [2.0, 9.0]


def test_dic_helpers_index_a_trace_read_back_by_name(tmp_path):
    from phylomap_b200 import api
    names = pb.colnames(pb.sumstatMCMC2sDICt)
    mat = np.arange(40, dtype=np.float64).reshape(4, 10)
    p = str(tmp_path / "t.rds")
    rds.write_rds(p, mat[:, ::-1].copy(), colnames=names[::-1])      # columns stored in another order
    back = rds.read_rds(p)
    np.testing.assert_array_equal(api._col(back, "l01", 6), mat[:, 6])
    np.testing.assert_array_equal(api._col(back, "log(p(y|Q))", -1), mat[:, 9])
    np.testing.assert_array_equal(api._col(mat, "l01", 6), mat[:, 6])


@pytest.mark.parametrize("version", [2, 3])
@pytest.mark.parametrize("compress", [True, "bzip2", "xz", False])
def test_formats_and_compressions(tmp_path, version, compress):
    obj = {"x": np.array([1.5, -2.0]), "names": ["a", "b"], "m": np.arange(6, dtype=np.int32).reshape(2, 3)}
    p = str(tmp_path / "f.rds")
    rds.write_rds(p, obj, compress=compress, version=version)
    back = rds.read_rds(p)
    np.testing.assert_array_equal(back["x"], obj["x"])
    np.testing.assert_array_equal(back["m"], obj["m"])
    assert back["names"] == ["a", "b"]


def test_rejects_what_it_cannot_read(tmp_path):
    p = str(tmp_path / "bad.rds")
    open(p, "wb").write(b"RDX2\nnot an rds stream")
    with pytest.raises(ValueError):
        rds.read_rds(p)
    import gzip
    open(p, "wb").write(gzip.compress(b"X\n" + (2).to_bytes(4, "big") + bytes(8) + (238).to_bytes(4, "big")))   # an ALTREP item
    with pytest.raises(ValueError):
        rds.read_rds(p)


@pytest.mark.skipif(not os.path.exists(REF_TREE), reason="the reference's data files are only present in the build container")
def test_make_squamate_tree_reproduces_the_shipped_fixture():
    """R/Squamate_tree_setup.R restated (synth.make_squamate_tree) from the newick file and the trait table gives the
    tree the package ships as .RData: ape's node numbering, tip labels, the 100-segment maps, the script's tip-state
    coding (trait "1" -> 2; the trait table's "2" falls through to the placeholder -10, as in the shipped file)."""
    from phylomap_b200 import synth
    d = os.path.dirname(REF_TREE)
    t = synth.make_squamate_tree(os.path.join(d, "squamate.phy"), os.path.join(d, "squamate_tipdata.csv"))
    z = pb.PhyloTree.read_rds(REF_TREE)
    np.testing.assert_array_equal(t.edge, z.edge)                       # read_newick numbers nodes like ape::read.tree
    np.testing.assert_allclose(t.edge_length, z.edge_length, rtol=1e-14)
    assert t.tip_label == list(z.tip_label)
    np.testing.assert_array_equal(t.states, z.states)
    for a, b in zip(t.mapnames, z.mapnames):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_allclose(np.concatenate(t.maps), np.concatenate(z.maps), rtol=1e-13)

"""Writes tests/golden/squamate_tree.npz from the reference's own fixture
inst/extdata/Squamate/phylomap_compatible_squamate_tree.RData (the tree the DIC vignette reads with readRDS,
vignettes/Squamate_DIC_model_selection.Rnw:78): 3 951 tips, 7 900 branches cut into 100 segments each
(R/Squamate_tree_setup.R:54-82) and the observed tip states.  /root/reference does not exist on the GPU box, so the GPU
tests read this derived copy.

    python tests/golden/make_squamate_fixture.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from phylomap_b200.tree import PhyloTree  # noqa: E402

SRC = "/root/reference/inst/extdata/Squamate/phylomap_compatible_squamate_tree.RData"
z = PhyloTree.read_rds(SRC)
off = np.zeros(z.E + 1, dtype=np.int64)
off[1:] = np.cumsum([len(m) for m in z.maps])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "squamate_tree.npz"), edge=z.edge, edge_length=z.edge_length,
                    states=z.states.astype(np.int8), maps_off=off, maps_len=np.concatenate(z.maps),
                    maps_state=np.concatenate(z.mapnames).astype(np.int8))
print("tips", z.T, "branches", z.E, "segments", int(off[-1]), "tree length", z.edge_length.sum())

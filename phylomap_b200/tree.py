"""Tree container in the reference's "tweaked phytools" layout and its flattening for the C ABI.

The reference reads these fields of the R list `x` (src/phylomap.cpp:896-910): `maps` (list of E numeric vectors),
`mapnames` (list of E integer vectors, 1-based states), `edge` (E x 2, parent/child, 1-based, tips 1..T),
`Nnode`, `node.states`, `states` (tip states, 1-based), and `edge.length` (EXP / DIC only).  Here `states` may also
be an [S, T] array: S independent sites on the same tree (the data-parallel axis the reference does not have).
"""
import numpy as np

from . import capi


class PhyloTree:
    def __init__(self, edge, edge_length, states=None, maps=None, mapnames=None, tip_label=None):
        self.edge = np.ascontiguousarray(edge, dtype=np.int32).reshape(-1, 2)
        self.edge_length = np.ascontiguousarray(edge_length, dtype=np.float64)
        self.E = self.edge.shape[0]
        self.T = (self.E + 2) // 2
        self.Nnode = self.T - 1
        self.states = None if states is None else np.asarray(states)
        if maps is None:  # one segment per branch, state 1 (phylomap_tutorial.Rnw:172-190)
            maps = [np.array([t]) for t in self.edge_length]
            mapnames = [np.array([1], dtype=np.int32) for _ in range(self.E)]
        self.maps = [np.asarray(m, dtype=np.float64) for m in maps]
        self.mapnames = [np.asarray(m, dtype=np.int32) for m in mapnames]
        self.tip_label = tip_label
        self._order = None

    # dict-style access with the R field names, so code reads like the vignettes (z$edge -> z["edge"])
    def __getitem__(self, k):
        return {"edge": self.edge, "edge.length": self.edge_length, "Nnode": self.Nnode, "states": self.states,
                "maps": self.maps, "mapnames": self.mapnames}[k]

    @classmethod
    def from_mapping(cls, z):
        """From the R list itself: a dict with the R field names, e.g. what `rds.read_rds` returns for a tree saved with
        saveRDS (inst/extdata/Squamate/phylomap_compatible_squamate_tree.RData)."""
        if isinstance(z, cls):
            return z
        from .rds import plain
        z = plain(z)
        states = z.get("states")
        if states is not None:
            states = np.asarray(states).astype(np.int32)  # R hands them over as doubles
        return cls(z["edge"], z.get("edge.length", [float(np.sum(m)) for m in z["maps"]]), states,
                   z.get("maps"), z.get("mapnames"), z.get("tip.label"))

    @classmethod
    def read_rds(cls, path):
        """readRDS(path) of a phylomap-compatible tree."""
        from .rds import read_rds
        return cls.from_mapping(read_rds(path))

    def to_mapping(self):
        """The R list (for `rds.write_rds`): field names and 1-based conventions of the reference."""
        out = {"edge": self.edge, "Nnode": np.array([self.Nnode], dtype=np.int32), "edge.length": self.edge_length,
               "maps": [m for m in self.maps], "mapnames": [m for m in self.mapnames]}
        if self.tip_label is not None:
            out["tip.label"] = list(self.tip_label)
        if self.states is not None:
            out["states"] = np.asarray(self.states, dtype=np.float64)
        from .rds import RObject
        return RObject(out, {"class": ["phylo"], "order": ["cladewise"]})

    def with_states(self, states, halve_tip_branches=True, segments=None):
        """Attach tip data the way simulate_2_state_tree does (R/simulate_2_state_tree.R:16-30): every tip branch
        becomes two halves named (1, tip state); with an [S, T] matrix the shared initial segmentation uses
        site 0's tip states (the segment states are redrawn by the first sweep anyway).
        segments=k instead cuts EVERY branch into k equal pieces (R/Squamate_tree_setup.R:54-82 uses 100): the first
        sweep needs B^(m-1) to connect the node states, which a one-segment branch cannot do for a sparse Q."""
        st = np.asarray(states)
        first = st if st.ndim == 1 else st[0]
        maps, names = list(self.maps), list(self.mapnames)
        if segments is not None:
            maps = [np.full(segments, t / segments) for t in self.edge_length]
            names = [np.ones(segments, dtype=np.int32) for _ in range(self.E)]
        elif halve_tip_branches:
            for e in range(self.E):
                c = self.edge[e, 1]
                if c <= self.T:
                    maps[e] = np.array([self.edge_length[e] / 2, self.edge_length[e] / 2])
                    names[e] = np.array([1, int(first[c - 1])], dtype=np.int32)
        return PhyloTree(self.edge, self.edge_length, st, maps, names, self.tip_label)

    def order(self):
        """(nen, nodelist, root): O(E) replacement of pruningwiseedgeorder / makenodelist / myreorder
        (R/sumstatMCMC.R:1-18), computed by the C ABI's pm_tree_order."""
        if self._order is None:
            import ctypes as C
            edge = np.asfortranarray(self.edge, dtype=np.int32)
            nen = np.zeros(self.E, dtype=np.int32)
            nodelist = np.zeros(max(self.T - 2, 1), dtype=np.int32)
            root = C.c_int32(0)
            err = C.create_string_buffer(512)
            rc = capi.lib().pm_tree_order(capi.ptr(edge), self.E, self.T, capi.ptr(nen), capi.ptr(nodelist),
                                          C.byref(root), err, 512)
            capi.check(rc, err)
            self._order = (nen, nodelist[:self.T - 2], int(root.value))
        return self._order

    def n_sites(self):
        return 1 if self.states.ndim == 1 else self.states.shape[0]

    def flat(self, nen=None, nodelist=None, root=None):
        """numpy arrays in the layout `pm_tree` wants; returns (struct, keepalive list)."""
        if nen is None:
            nen, nodelist, root = self.order()
        edge = np.asfortranarray(self.edge, dtype=np.int32)
        nen = np.ascontiguousarray(nen, dtype=np.int32)
        nodelist = np.ascontiguousarray(nodelist, dtype=np.int32)
        off = np.zeros(self.E + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(m) for m in self.maps])
        mlen = np.ascontiguousarray(np.concatenate(self.maps), dtype=np.float64)
        mst = np.ascontiguousarray(np.concatenate(self.mapnames), dtype=np.int32)
        st = self.states
        if st is None:
            raise ValueError("tree has no tip states")
        st = st[None, :] if st.ndim == 1 else st
        if st.shape[1] != self.T:
            raise ValueError("states must have one entry per tip")
        t = capi.PmTree()
        keep = [edge, nen, nodelist, off, mlen, mst]
        if st.dtype == np.uint8:
            st8 = np.ascontiguousarray(st)
            keep.append(st8)
            t.states, t.states_u8 = None, capi.ptr(st8)
        else:
            st32 = np.ascontiguousarray(st, dtype=np.int32)
            keep.append(st32)
            t.states, t.states_u8 = capi.ptr(st32), None
        t.n_tips, t.n_edges = self.T, self.E
        t.edge, t.nen, t.nodelist, t.root = capi.ptr(edge), capi.ptr(nen), capi.ptr(nodelist), int(root)
        t.maps_off, t.maps_len, t.maps_state = capi.ptr(off), capi.ptr(mlen), capi.ptr(mst)
        t.n_sites = st.shape[0]
        el = np.ascontiguousarray(self.edge_length, dtype=np.float64)
        keep.append(el)
        t.edge_length = capi.ptr(el)
        return t, keep

    def oracle_dict(self, nen=None, nodelist=None, root=None):
        """The same tree as the dict oracle/bridge.py takes (tests only)."""
        if nen is None:
            nen, nodelist, root = self.order()
        off = np.zeros(self.E + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(m) for m in self.maps])
        st = self.states
        st = st[None, :] if st.ndim == 1 else st
        return {"edge": self.edge, "nen": nen, "nodelist": nodelist, "root": root, "maps_off": off,
                "maps_len": np.concatenate(self.maps), "maps_state": np.concatenate(self.mapnames),
                "states": np.ascontiguousarray(st, dtype=np.int32), "edge_length": self.edge_length}
